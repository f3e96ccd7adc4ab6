"""ORACLE / TEST INFRASTRUCTURE ONLY.  ``greensim.tags`` as consumed by ``itsim/__init__.py:5-6,9-15``."""
from enum import Enum
from typing import Iterator, Set


class Tags(Enum):
    """Base class for tag enumerations (``class Tag(Tags)`` in ``itsim/__init__.py:9``)."""
    pass


class TaggedObject:
    """Carries a set of tags fixed at construction (pinned by ``tests/it_objects/test_it_object.py:4-12``)."""

    def __init__(self, *tag_set: Tags) -> None:
        super().__init__()
        self._tag_set: Set[Tags] = set(tag_set)

    def iter_tags(self) -> Iterator[Tags]:
        return iter(self._tag_set)

    def has_tag(self, needle: Tags) -> bool:
        return needle in self._tag_set

    def tag_with(self, *tag_set: Tags) -> None:
        self._tag_set |= set(tag_set)

    def untag(self, *tag_set: Tags) -> None:
        self._tag_set -= set(tag_set)

    def clear_tags(self) -> None:
        self._tag_set.clear()
