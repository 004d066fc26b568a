"""``python -m pylabolt_b200 --solver fluidLB`` = the reference's ``pylabolt``
console script (setup.cfg entry point -> pylabolt/pylabolt.py:main)."""
from .cli import main

main()
