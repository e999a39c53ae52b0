"""Key selection helpers -- mirror of /root/reference/src/utils/state_dict.py:8-42 (part of the drop-in boundary)."""
from __future__ import annotations

import re
from typing import Sequence

from pydantic import BaseModel


class RegexMatch(BaseModel):
    regex: str

    def __call__(self, value: str) -> bool:
        return re.match(self.regex, value) is not None


def get_target_keys(include: Sequence[str | RegexMatch], exclude: Sequence[str | RegexMatch], keys: list[str]) -> list[str]:
    """Keys selected by `include` (substring or anchored regex) minus those hit by `exclude`."""

    def hits(pattern, key: str) -> bool:
        if isinstance(pattern, str):
            return pattern in key
        if isinstance(pattern, RegexMatch):
            return re.compile(pattern.regex).match(key) is not None
        return False

    chosen = {k for k in keys if any(hits(p, k) for p in include)}
    chosen -= {k for k in keys if any(hits(p, k) for p in exclude)}
    return list(chosen)
