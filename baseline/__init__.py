"""baseline package of sparkflow_b200."""
