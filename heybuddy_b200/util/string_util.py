"""``safe_name`` -- file naming of the feature cache (reference util/string_util.py:145-151)."""
import re

__all__ = ["safe_name"]


def safe_name(name: str) -> str:
    """Lower-case, every non-alphanumeric run -> one underscore, stripped of edge underscores."""
    return re.sub(r"_+", "_", re.sub(r"[^a-z0-9]", "_", name.lower())).strip("_")
