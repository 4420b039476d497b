"""B200-native batched simulator for the Ackermann env-step hot path of ulusoyn/mujoco_playground."""
__all__ = ["BatchedAckermannEnv", "AckermannRobotEnv", "AckermannGymnasiumMazeEnv", "AckermannB200VecEnv", "make_ackermann_env",
           "list_available_mazes"]


def __getattr__(name):
    if name in ("BatchedAckermannEnv", "AckermannRobotEnv", "AckermannGymnasiumMazeEnv"):
        from . import env
        return getattr(env, name)
    if name == "AckermannB200VecEnv":
        from . import vec_env
        return vec_env.AckermannB200VecEnv
    if name in ("make_ackermann_env", "list_available_mazes"):
        from . import make_env
        return getattr(make_env, name)
    raise AttributeError(name)
