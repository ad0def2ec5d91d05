"""CPU oracles for the Tetris placement hot path — TEST INFRASTRUCTURE ONLY (see pyoracle.py, coracle.c)."""
