"""ngsepcore_b200 — B200-native pileup / SNVQ hot path of NGSEP 4.3.2."""
