"""Runnable counterparts of the reference's example scripts (cinema/examples/train)."""
