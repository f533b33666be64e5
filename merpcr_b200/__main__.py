"""Entry point: python -m merpcr_b200 (mirror of the reference's merpcr/__main__.py)."""
from .cli import main

if __name__ == "__main__":
    raise SystemExit(main())
