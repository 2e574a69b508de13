"""Drop-in import shim: `from src.flows import ...` / `from src.models import ...` resolve to the B200
implementation with the reference's names (the reference is importable as `src.flows` / `src.models`,
pyproject.toml:37-38)."""
