"""Reference-facing package: same module names as the reference's `src` (src/__init__.py:1-4) so that
`python -m src.main <config.yaml> [--test]` and the YAML name registry resolve to the B200 implementation."""
from . import data, model, runner, callbacks  # noqa: F401
