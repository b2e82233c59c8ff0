"""graph package of sparkflow_b200."""
