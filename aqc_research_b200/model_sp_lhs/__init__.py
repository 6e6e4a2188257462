"""B200-native drop-ins for the reference's model_sp_lhs package (host-side mirror of the API; compute lives in libaqc_b200.so)."""
