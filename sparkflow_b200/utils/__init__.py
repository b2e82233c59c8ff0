"""utils package of sparkflow_b200."""
