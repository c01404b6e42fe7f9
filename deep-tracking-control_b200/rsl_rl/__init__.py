"""Mirror of the reference's `rsl_rl` module paths for the hot path (SURVEY.md section 8b surface 2)."""
