"""Environment factory: the counterpart of the reference's ``rl.make_env`` (shipped only as
``src/rl/__pycache__/make_env.cpython-312.pyc``; signature, defaults and messages recovered from its constants).

    make_ackermann_env(env_type='maze', maze_id='PointMaze_UMaze-v3', render_mode=None, max_linear_velocity=0.5,
                       max_angular_velocity=1.0, goal_distance_threshold=0.3, **kwargs)
    list_available_mazes()

Note the factory's defaults (0.5 m/s, 0.3 m) differ from the class defaults (1.0 m/s, 0.5 m; ackermann_env.py:51-60).
The PointMaze scenes need the maze XML of the un-vendored ``gymnasium_robotics`` package: as in the reference
when that package is missing, ``env_type='maze'`` prints the reference's warning and falls back to the simple environment.
``num_envs`` (new) returns the batched device environment instead of the single-environment gym adapter.
"""
from __future__ import annotations

MAZE_IDS = ("PointMaze_UMaze-v3", "PointMaze-Open-v3", "PointMaze-Medium-v3", "PointMaze-Large-v3")


def _has_gymnasium_maze() -> bool:
    try:
        import gymnasium_robotics  # noqa: F401
        return True
    except Exception:
        return False


def list_available_mazes():
    """List available Gymnasium Robotics maze environments (empty when the package is absent, like the reference)."""
    if not _has_gymnasium_maze():
        print("gymnasium-robotics not installed. No maze environments available.")
        return []
    return list(MAZE_IDS)


def make_ackermann_env(env_type: str = "maze", maze_id: str = "PointMaze_UMaze-v3", render_mode=None, max_linear_velocity: float = 0.5,
                       max_angular_velocity: float = 1.0, goal_distance_threshold: float = 0.3, num_envs: int = 0, **kwargs):
    from .env import AckermannRobotEnv, BatchedAckermannEnv
    if env_type == "maze":
        # the maze scenes are not built (see module docstring); same fallback path as the reference without gymnasium-robotics
        print("Warning: gymnasium-robotics not installed. Falling back to simple environment.")
        print("Install with: pip install gymnasium-robotics")
        env_type = "simple"
    if env_type != "simple":
        raise ValueError(f"Unknown environment type: {env_type}. Use 'maze' or 'simple'.")
    common = dict(render_mode=render_mode, max_linear_velocity=max_linear_velocity, max_angular_velocity=max_angular_velocity,
                  goal_distance_threshold=goal_distance_threshold, **kwargs)
    if num_envs and num_envs > 0:
        return BatchedAckermannEnv(num_envs, **common)
    return AckermannRobotEnv(**common)
