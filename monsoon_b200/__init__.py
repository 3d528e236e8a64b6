"""monsoon_b200 -- B200-native batched Stormbound simulator + heuristic-agent evaluator.

Drop-in for the data-parallel hot path of dvrp0/Monsoon's evolutionary training
(games/abstract_game.py, evo/game_adapter.py, evo/fitness.py); see DESIGN.md and INTEGRATION.md.
Importing the package does not need a GPU; constructing an Engine does (no CPU fallback).
"""
from ._card_table import CARDS  # noqa: F401

__all__ = ["CARDS"]
