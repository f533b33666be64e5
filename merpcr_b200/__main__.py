"""`python -m merpcr_b200 ...` == the `merpcr` console script of the reference (src/merpcr/__main__.py)."""
from merpcr_b200.cli import main as _cli_main

raise SystemExit(_cli_main())
