#!/usr/bin/env python
"""Reference-compatible entry point: ``python launch.py -c config.py [-p port]`` (see
skycomputing_b200/launch.py; rank discovery supports srun, torchrun and ``--spawn N``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.launch import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
