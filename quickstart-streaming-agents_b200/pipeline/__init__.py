"""The Lab2 topic graph as a local consumer/producer loop (replaces the Flink statements of
terraform/lab2-vector-search/main.tf:233-331)."""
