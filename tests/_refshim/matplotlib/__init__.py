"""Import-time stub (train_boxpose.py:31, internal/vis.py:20); plotting is not on the hot path."""
