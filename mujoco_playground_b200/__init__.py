"""B200-native batched simulator for the Ackermann env-step hot path of ulusoyn/mujoco_playground."""
__all__ = ["BatchedAckermannEnv", "AckermannRobotEnv"]


def __getattr__(name):
    if name in __all__:
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
