"""A small LRU of captured CUDA graphs.  Every captured graph pins a private memory pool holding the full activation
set of its shape (hundreds of MB for the VAE at 512x512); the reference application varies H / W per request
(image_resolution and the aspect ratio, editany_lora.py:760-769), so an unbounded per-shape cache would grow GPU
memory for the lifetime of the process.  Evicting an entry drops the graph and its static tensors (its pool is
returned to the allocator)."""
from collections import OrderedDict


class GraphLRU:
    def __init__(self, capacity=4):
        self.capacity, self._d = capacity, OrderedDict()

    def get(self, key):
        v = self._d.get(key)
        if v is not None:
            self._d.move_to_end(key)
        return v

    def put(self, key, value):
        self._d[key] = value
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return value

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        return key in self._d
