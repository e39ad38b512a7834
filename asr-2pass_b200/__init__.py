"""B200-native offline Paraformer acoustic-model path."""
