"""Import-time stand-in: the reference's utils/det_util.py imports matplotlib for plotting helpers that the hot path
never calls (oracle/shims/README.md)."""
