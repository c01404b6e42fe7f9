"""Mirror of the reference's `legged_gym` module paths for the hot path (SURVEY.md section 8b surface 1)."""
