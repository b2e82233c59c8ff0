"""io package of sparkflow_b200."""
