from .actor_critic_decoder import ActorCriticDecoder, AC_Args, reference_init_state_dict  # noqa: F401
from .memory import Memory  # noqa: F401
