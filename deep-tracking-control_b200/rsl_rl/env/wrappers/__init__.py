from .history_wrapper import HistoryWrapper  # noqa: F401
