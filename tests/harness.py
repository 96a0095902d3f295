"""Re-export of the A/B main-loop harness (lives in oracle/ because bench.py's reference arm uses it too)."""
from oracle.mainloop import MainLoop, pitch, ptr, to_dev, to_host  # noqa: F401
