"""models package of sparkflow_b200."""
