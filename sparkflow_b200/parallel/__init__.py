"""parallel package of sparkflow_b200."""
