"""TEST INFRASTRUCTURE -- minimal stand-in for utiasASRL/pysteam (un-vendored, un-pinned;
reference README.md:39-44). Restates the parts of the STEAM evaluator graph / Gauss-Newton
machinery that reference mpc.py:8-16,88-146 and pysteam_augmented/*.py use. PARITY UNPINNED."""
