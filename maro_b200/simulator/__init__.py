from .env import DecisionMode, Env

__all__ = ["Env", "DecisionMode"]
