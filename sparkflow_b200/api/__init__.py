"""api package of sparkflow_b200."""
