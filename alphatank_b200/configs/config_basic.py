"""Mirror of the reference's configs/config_basic.py (same names, same values), minus the pygame key
bindings of the human-play modes, which the batched env does not support.

The physics / reward constants are compiled into the CUDA kernels (alphatank_b200/csrc/at_sim.h cites
the lines); they are listed here so code written against the reference's ``from configs.config_basic
import *`` keeps working.  The three switches that are runtime options of the batched env are
USE_OCTAGON, BUFF_ON / DEBUFF_ON and TERMINATE_TIME: the env constructors read them from this module
unless overridden by keyword.
"""
WIDTH, HEIGHT = 770, 770
MAZEWIDTH, MAZEHEIGHT = 11, 11
GRID_SIZE = WIDTH / MAZEWIDTH

WHITE = (255, 255, 255)
BLACK = (0, 0, 0)
GREEN = (0, 255, 0)
RED = (255, 0, 0)
GRAY = (100, 100, 100)
YELLOW = (255, 255, 0)

# ----------------- GAME SETTING (configs/config_basic.py:19-37)
EPSILON = 0.01
ROTATION_SPEED = 2
BULLET_SPEED = 1
BULLET_MAX_BOUNCES = 6
BULLET_MAX_DISTANCE = 300
MAX_BULLETS = 6
BULLET_COOLDOWN = 300
STATIONARY_EPSILON = 3
USE_OCTAGON = True
ROTATION_DEGREE = 3
TANK_SPEED = 10
TANK_WIDTH = 30
TANK_HEIGHT = 24

two_tank_configs = {
    "Tank1": {"team": "TeamA", "color": GREEN, "mode": "agent"},
    "Tank2": {"team": "TeamB", "color": RED, "mode": "agent"},
}

# ----------------- REWARD CONFIG (configs/config_basic.py:70-119)
HIT_PENALTY = -50
TEAM_HIT_PENALTY = -5
OPPONENT_HIT_REWARD = 50
VICTORY_REWARD = 200
WALL_HIT_THRESHOLD = 0
WALL_HIT_STRONG_PENALTY = 0
WALL_HIT_PENALTY = 0
STATIONARY_PENALTY = 0
MOVE_REWARD = 0
BFS_FORWARD_REWARD = 0
BFS_BACKWARD_PENALTY = 0
BFS_PATH_LEN_REWARD = 0
BFS_PATH_LEN_PENALTY = 0
TRAJECTORY_HIT_REWARD = 40
TRAJECTORY_DIST_REWARD = 10
TRAJECTORY_DIST_PENALTY = -10
TRAJECTORY_FAR_THRESHOLD = 300
TRAJECTORY_DIST_THRESHOLD = 200
DODGE_FACTOR = 0.01
TRAJECTORY_AIM_REWARD = 0
AIMING_FRAMES_THRESHOLD = 17
BULLET_AWAY_PENALTY = 0.02
BULLET_CLOSE_REWARD = 0.01
DISTANCE_CANCEL_THRESHOLD = 100
ACTION_CONSISTENCY_REWARD = 0
ACTION_CHANGE_PENALTY = 0
ROTATION_PENALTY = 0
ROTATION_THRESHOLD = 20
ROTATION_RESET_DISTANCE = 50
CONTROL_CHANGE_PENALTY = 0
CONTROL_CHANGE_THRESHOLD = 0.5

# Buff & Debuff (configs/config_basic.py:123-124)
BUFF_ON = True
DEBUFF_ON = True

# Set 512 if infinite life, None in inference (configs/config_basic.py:127)
TERMINATE_TIME = None

VISUALIZE_TRAJ = False
RENDER_AIMING = True
RENDER_BFS = True
VISUALIZE_EXPLOSION = True
