from .strategy_searcher import StrategySearcher  # noqa: F401
