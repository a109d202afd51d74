"""Ray actors used by the declarative group API (only importable inside a Ray installation).

`Info` plays the role of the reference's detached metadata actor
(python/ray/util/collective/util.py:54-84): the driver stores (actor ids, world size, ranks,
backend, timeout) once; each member actor reads it on its first collective call.
"""
import ray


@ray.remote
class Info:
    def __init__(self):
        self.ids = None
        self.world_size = -1
        self.rank = -1
        self.backend = None
        self.gloo_timeout = 30000

    def set_info(self, ids, world_size, rank, backend, gloo_timeout):
        self.ids, self.world_size, self.rank = ids, world_size, rank
        self.backend, self.gloo_timeout = backend, gloo_timeout

    def get_info(self):
        return self.ids, self.world_size, self.rank, self.backend, self.gloo_timeout
