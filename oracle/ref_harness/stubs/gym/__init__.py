"""Import stub for OpenAI gym: the reference only subclasses gym.Wrapper (history_wrapper.py:6)."""


class Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def reset_idx(self, env_ids):
        return self.env.reset_idx(env_ids)

    def step(self, action):
        return self.env.step(action)
