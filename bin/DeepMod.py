#!/usr/bin/env python
"""Same entry point name as the reference's bin/DeepMod.py; only `detect` is provided."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepmod_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
