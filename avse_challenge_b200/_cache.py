"""Bounded per-shape caches (no torch dependency: also used by the host-only drivers)."""
from __future__ import annotations

from collections import OrderedDict

# Per-shape plans (activation workspace, CUDA graph) kept per engine.  The reference's evaluation flow runs batch 1 with
# every utterance a different length (train_wsj0mix.py:503-604), so the caches are LRU-bounded: an unbounded dict would
# hold one full workspace + one graph pool per distinct length until the device runs out of memory.
SHAPE_CACHE_ENTRIES = 8


class LRUDict(OrderedDict):
    """``dict`` with at most ``cap`` entries; reads refresh an entry, inserts evict the least recently used one
    (``on_evict(key, value)`` lets the owner drop whatever was captured against the evicted buffers)."""

    def __init__(self, cap: int = SHAPE_CACHE_ENTRIES, on_evict=None):
        super().__init__()
        self.cap, self.on_evict = max(1, int(cap)), on_evict

    def get(self, key, default=None):
        if key in self:
            self.move_to_end(key)
            return super().__getitem__(key)
        return default

    def __getitem__(self, key):
        v = super().__getitem__(key)
        self.move_to_end(key)
        return v

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.move_to_end(key)
        while len(self) > self.cap:
            k, v = self.popitem(last=False)
            if self.on_evict is not None:
                self.on_evict(k, v)
