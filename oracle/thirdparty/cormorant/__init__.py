"""ORACLE / TEST INFRASTRUCTURE — restated stand-in for risilab/cormorant @6a4b6370 (requirements.txt:3).
Not the upstream package: parity unpinned (see oracle/README.md)."""
