"""``python -m pyflow ...`` is the ``pyflow`` console command (see ``pyflow/cli.py``)."""
import sys

from pyflow.cli import main

sys.exit(main(sys.argv[1:]))
