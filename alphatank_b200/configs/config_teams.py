"""Mirror of the reference's configs/config_teams.py: the same dictionaries (names, teams, colours,
modes, bot types; configs/config_teams.py:5-227) without the pygame key bindings.  ``mixed_team_configs``
contains ``mode: "human"`` tanks and is listed for completeness; MultiAgentTeamEnv rejects it."""
from .config_basic import *  # noqa: F401,F403
from .config_basic import GRAY, GREEN, RED


def _t(team, color, mode, bot_type=None):
    d = {"team": team, "color": color, "mode": mode}
    if bot_type is not None:
        d["bot_type"] = bot_type
    return d


team_configs = {
    "Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamA", RED, "agent"),
    "Tank3": _t("TeamB", GREEN, "agent"), "Tank4": _t("TeamB", GREEN, "agent"),
}
pair_configs = {"Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamB", GREEN, "agent")}
team_vs_bot_configs = {
    "Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamA", RED, "agent"),
    "Tank3": _t("TeamB", GREEN, "bot", "smart"), "Tank4": _t("TeamB", GREEN, "bot", "smart"),
}
team_vs_bot_hard_configs = {
    "Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamA", RED, "agent"),
    "Tank3": _t("TeamB", GREEN, "bot", "aggressive"), "Tank4": _t("TeamB", GREEN, "bot", "smart"),
    "Tank5": _t("TeamB", GREEN, "bot", "defensive"),
}
team_vs_bot_harder_configs = {
    "Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamA", RED, "agent"),
    "Tank3": _t("TeamB", GREEN, "bot", "aggressive"), "Tank4": _t("TeamB", GREEN, "bot", "smart"),
    "Tank5": _t("TeamB", GREEN, "bot", "defensive"), "Tank6": _t("TeamB", GREEN, "bot", "smart"),
}
team_vs_def_bot_configs = {
    "Tank1": _t("TeamA", RED, "agent"), "Tank2": _t("TeamA", RED, "agent"),
    "Tank3": _t("TeamB", GREEN, "bot", "defensive"),
}
bot_team_configs = {
    "Tank1": _t("TeamA", GREEN, "bot", "smart"), "Tank2": _t("TeamA", GREEN, "bot", "smart"),
    "Tank3": _t("TeamB", RED, "bot", "aggressive"), "Tank4": _t("TeamB", RED, "bot", "aggressive"),
    "Tank5": _t("TeamC", GRAY, "bot", "defensive"), "Tank6": _t("TeamC", GRAY, "bot", "defensive"),
}
mixed_team_configs = {
    "Tank1": _t("TeamA", GREEN, "human"), "Tank2": _t("TeamA", GREEN, "bot", "smart"),
    "Tank3": _t("TeamA", GREEN, "agent"), "Tank4": _t("TeamB", RED, "human"),
    "Tank5": _t("TeamB", RED, "bot", "smart"), "Tank6": _t("TeamB", RED, "agent"),
    "Tank7": _t("TeamC", GRAY, "bot", "aggressive"), "Tank8": _t("TeamC", GRAY, "bot", "smart"),
}

# Need to be consistent with training (configs/config_teams.py:232)
inference_agent_configs = {"Tank1": "checkpoints/team_ppo/ppo_agent_0.pt", "Tank2": "checkpoints/team_ppo/ppo_agent_1.pt"}
