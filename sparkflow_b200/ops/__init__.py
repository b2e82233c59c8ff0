"""ops package of sparkflow_b200."""
