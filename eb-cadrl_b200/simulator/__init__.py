"""`simulator` — the reference's package name, served by the B200 hot path (simulator/__init__.py:1-7)."""
from ebc import gymlite

gymlite.register(id="EntityBasedCollisionAvoidance-v0", entry_point="simulator.env:EntityBasedCollisionAvoidance")
make = gymlite.make
