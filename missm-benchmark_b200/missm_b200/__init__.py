"""missm_b200 -- host-side Python of the B200-native MissM-Benchmark hot path."""
