"""Import stub: the reference uses PrefixProto only as a base class (actor_critic_decoder.py:11)."""


class PrefixProto:
    def __init_subclass__(cls, cli=False, **kw):
        super().__init_subclass__(**kw)
