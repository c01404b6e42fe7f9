class SubTerrain:
    def __init__(self, *a, **k):
        raise RuntimeError("isaacgym stub: terrain generation is out of scope (SURVEY.md section 8f N3)")
