"""A 100-line stand-in for Ray actors so the reference's collective tests port almost verbatim.

    actors = [spawn(Worker, env={...}) for _ in range(2)]
    results = get([a.do_allreduce.remote() for a in actors])     # ~ ray.get([...])

Each actor is a subprocess that instantiates the class and executes method calls in order.
Exceptions raised remotely are re-raised by get() with their original type (the reference tests
assert on RuntimeError / ValueError).  CPU tests fork (fast); GPU tests must spawn (CUDA).
"""
import multiprocessing as mp
import os
import pickle
import traceback


class RemoteError(RuntimeError):
    pass


def _actor_main(conn, cls, env, args, kwargs):
    os.environ.update(env)
    try:
        obj = cls(*args, **kwargs)
        conn.send(("ok", None))
    except BaseException as e:  # noqa: BLE001
        conn.send(("err", (type(e).__name__, str(e), traceback.format_exc(), _try_pickle(e))))
        return
    while True:
        try:
            msg = conn.recv()
        except EOFError:
            return
        if msg is None:
            return
        call_id, method, a, kw = msg
        try:
            if method == "__ray_call__":  # run an arbitrary function against the actor instance
                import cloudpickle

                fn = cloudpickle.loads(a[0])
                conn.send((call_id, "ok", fn(obj, *a[1:], **kw)))
                continue
            conn.send((call_id, "ok", getattr(obj, method)(*a, **kw)))
        except BaseException as e:  # noqa: BLE001
            conn.send((call_id, "err", (type(e).__name__, str(e), traceback.format_exc(), _try_pickle(e))))


def _try_pickle(e):
    try:
        return pickle.dumps(e)
    except Exception:
        return None


class _Ref:
    def __init__(self, actor, call_id):
        self.actor, self.call_id = actor, call_id


class _Method:
    def __init__(self, actor, name):
        self.actor, self.name = actor, name

    def remote(self, *args, **kwargs):
        return self.actor._submit(self.name, args, kwargs)


class _RayCall:
    def __init__(self, actor):
        self.actor = actor

    def remote(self, fn, *args, **kwargs):
        import cloudpickle

        return self.actor._submit("__ray_call__", (cloudpickle.dumps(fn),) + args, kwargs)


class ActorId:
    """What travels when an actor handle is pickled into another actor (identity only)."""

    def __init__(self, uid):
        self._ray_actor_id = uid

    def __eq__(self, other):
        return getattr(other, "_ray_actor_id", None) == self._ray_actor_id

    def __hash__(self):
        return hash(self._ray_actor_id)


class Actor:
    _counter = 0

    def __init__(self, proc, conn):
        self._proc, self._conn = proc, conn
        self._next = 0
        self._done = {}
        Actor._counter += 1
        self._ray_actor_id = f"actor-{os.getpid()}-{Actor._counter}"

    def __reduce__(self):
        return (ActorId, (self._ray_actor_id,))

    def __eq__(self, other):
        return getattr(other, "_ray_actor_id", None) == self._ray_actor_id

    def __hash__(self):
        return hash(self._ray_actor_id)

    def __getattr__(self, name):
        if name == "__ray_call__":
            return _RayCall(self)
        if name.startswith("_"):
            raise AttributeError(name)
        return _Method(self, name)

    def _submit(self, name, args, kwargs):
        cid = self._next
        self._next += 1
        self._conn.send((cid, name, args, kwargs))
        return _Ref(self, cid)

    def _result(self, cid, timeout):
        while cid not in self._done:
            if not self._conn.poll(timeout):
                raise TimeoutError(f"actor call {cid} did not finish within {timeout}s")
            rid, status, payload = self._conn.recv()
            self._done[rid] = (status, payload)
        status, payload = self._done.pop(cid)
        if status == "ok":
            return payload
        name, msg, tb, pickled = payload
        if pickled is not None:
            try:
                raise pickle.loads(pickled)
            except (pickle.UnpicklingError, AttributeError, ImportError):
                pass
        import builtins

        exc_type = getattr(builtins, name, None)
        if isinstance(exc_type, type) and issubclass(exc_type, BaseException):
            raise exc_type(f"{msg}\n--- remote traceback ---\n{tb}")
        raise RemoteError(f"{name}: {msg}\n--- remote traceback ---\n{tb}")

    def kill(self):
        try:
            self._conn.send(None)
        except Exception:
            pass
        self._proc.join(timeout=5)
        if self._proc.is_alive():
            self._proc.terminate()
            self._proc.join(timeout=5)


def spawn(cls, *args, env=None, start_method="fork", **kwargs) -> Actor:
    ctx = mp.get_context(start_method)
    parent, child = ctx.Pipe()
    proc = ctx.Process(target=_actor_main, args=(child, cls, dict(env or {}), args, kwargs), daemon=True)
    proc.start()
    child.close()
    if not parent.poll(120):
        proc.terminate()
        raise TimeoutError("actor did not start")
    status, payload = parent.recv()
    if status != "ok":
        raise RemoteError(f"actor constructor failed: {payload[0]}: {payload[1]}\n{payload[2]}")
    return Actor(proc, parent)


def get(refs, timeout=120):
    """ray.get for one ref or a list of refs."""
    if isinstance(refs, _Ref):
        return refs.actor._result(refs.call_id, timeout)
    # drain in submission order but surface the first error only after every call has finished,
    # otherwise a failing rank would leave its peers' results unread
    out, first_err = [], None
    for r in refs:
        try:
            out.append(r.actor._result(r.call_id, timeout))
        except BaseException as e:  # noqa: BLE001
            out.append(None)
            first_err = first_err or e
    if first_err is not None:
        raise first_err
    return out
