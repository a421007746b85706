from .vector_env import VectorEnv

__all__ = ["VectorEnv"]
