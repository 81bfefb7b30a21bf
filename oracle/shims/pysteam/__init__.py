"""TEST INFRASTRUCTURE -- outer clone directory of the pysteam stand-in (reference README.md:39-44
clones pysteam into the repo root, hence the double `pysteam.pysteam` at mpc.py:8-15)."""
