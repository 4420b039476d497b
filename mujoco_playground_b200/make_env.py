"""Environment factory: the counterpart of the reference's ``rl.make_env`` (shipped only as
``src/rl/__pycache__/make_env.cpython-312.pyc``; signature, defaults and messages recovered from its constants).

    make_ackermann_env(env_type='maze', maze_id='PointMaze_UMaze-v3', render_mode=None, max_linear_velocity=0.5,
                       max_angular_velocity=1.0, goal_distance_threshold=0.3, **kwargs)
    list_available_mazes()

Note the factory's defaults (0.5 m/s, 0.3 m) differ from the class defaults (1.0 m/s, 0.5 m; ackermann_env.py:51-60).
The maze layouts are built in (compiler/maze.py restates the maps of the un-vendored ``gymnasium_robotics`` package), so the
reference's "gymnasium-robotics not installed" fallback is never taken.  ``num_envs`` (new) returns the batched device
environment instead of the single-environment gym adapter.
"""
from __future__ import annotations

MAZE_IDS = ("PointMaze_UMaze-v3", "PointMaze-Open-v3", "PointMaze-Medium-v3", "PointMaze-Large-v3")


def list_available_mazes():
    """List available maze environment ids (the four PointMaze layouts the reference documents)."""
    return list(MAZE_IDS)


def make_ackermann_env(env_type: str = "maze", maze_id: str = "PointMaze_UMaze-v3", render_mode=None, max_linear_velocity: float = 0.5,
                       max_angular_velocity: float = 1.0, goal_distance_threshold: float = 0.3, num_envs: int = 0, **kwargs):
    from .compiler.maze import MAZE_ENV_IDS
    from .env import AckermannGymnasiumMazeEnv, AckermannRobotEnv, BatchedAckermannEnv
    common = dict(render_mode=render_mode, max_linear_velocity=max_linear_velocity, max_angular_velocity=max_angular_velocity,
                  goal_distance_threshold=goal_distance_threshold, **kwargs)
    if env_type == "maze":
        if num_envs and num_envs > 0:
            if maze_id not in MAZE_ENV_IDS:
                raise ValueError(f"Failed to load maze environment '{maze_id}'")
            return BatchedAckermannEnv(num_envs, model="maze:" + MAZE_ENV_IDS[maze_id], **common)
        return AckermannGymnasiumMazeEnv(maze_env_id=maze_id, **common)
    if env_type == "simple":
        if num_envs and num_envs > 0:
            return BatchedAckermannEnv(num_envs, **common)
        return AckermannRobotEnv(**common)
    raise ValueError(f"Unknown environment type: {env_type}. Use 'maze' or 'simple'.")
