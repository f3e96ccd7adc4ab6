"""
Process tags, as ``itsim.Tag`` (``itsim/__init__.py:9-11``) and greensim's cascade: a process carries the tags of its
body (``@tagged(...)``, ``@malware``: ``itsim/__init__.py:52-57``) and those of the process that created it.  The engine
keeps them as a bit mask per process (bit = tag value) and stamps them on the telemetry records a process causes.
"""
from enum import IntEnum
from typing import Iterable, List


class Tag(IntEnum):
    MALWARE = 0
    VULNERABLE = 1


def tag_names(mask: int) -> List[str]:
    """``[t.name for t in process.iter_tags()]`` for a mask read from a record."""
    return [t.name for t in Tag if mask & (1 << t.value)]


def tag_mask(tags: Iterable[Tag]) -> int:
    mask = 0
    for t in tags:
        mask |= 1 << int(t)
    return mask
