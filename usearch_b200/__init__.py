"""B200-native batched HNSW search behind the USearch API."""
