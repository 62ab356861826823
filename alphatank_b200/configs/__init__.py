"""Mirrors of the reference's configs/ package (config_basic.py, config_teams.py) without pygame."""
