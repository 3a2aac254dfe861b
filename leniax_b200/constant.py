"""Constants of the hot path (reference: leniax/constant.py:6-13)."""
EPSILON = 1e-7
START_CHECK_STOP = 10
NB_STATS_STEPS = 128
NB_CHARS = (ord('Z') - ord('A')) + (ord('z') - ord('a')) + (ord('þ') - ord('À'))
