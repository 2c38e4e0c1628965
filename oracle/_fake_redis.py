"""Dict-backed stand-in for the ``redis`` package (test infrastructure).

Only used by oracle/gen_golden.py, in the build container, to run the REFERENCE's own
``RedisVectorStore`` (radiant/storage/redis_store.py) without a Redis server: hashes,
SCAN in insertion order, pipelines of HGETALL, and no RediSearch module - so the
reference takes its exact linear-scan path (``_retrieve_by_embedding_linear``).
Install with ``install()`` BEFORE importing radiant.storage.redis_store.
"""

from __future__ import annotations

import fnmatch
import sys
import types
from typing import Any, Dict, List, Optional


class ResponseError(Exception):
    pass


class ConnectionError(Exception):  # noqa: A001 - mirrors redis.ConnectionError
    pass


_DB: Dict[str, Dict[bytes, bytes]] = {}


def _b(x: Any) -> bytes:
    if isinstance(x, bytes):
        return x
    return str(x).encode("utf-8")


class _Pipeline:
    def __init__(self, r: "Redis") -> None:
        self._r = r
        self._ops: List[Any] = []

    def hgetall(self, key: str) -> "_Pipeline":
        self._ops.append(("hgetall", key))
        return self

    def hset(self, key: str, field: Optional[str] = None, value: Any = None, mapping=None) -> "_Pipeline":
        self._ops.append(("hset", key, field, value, mapping))
        return self

    def execute(self) -> List[Any]:
        out = []
        for op in self._ops:
            if op[0] == "hgetall":
                out.append(self._r.hgetall(op[1]))
            else:
                out.append(self._r.hset(op[1], op[2], op[3], op[4]))
        self._ops = []
        return out


class Redis:
    def __init__(self, decode_responses: bool = False) -> None:
        self._decode = decode_responses

    @classmethod
    def from_url(cls, url: str, decode_responses: bool = False) -> "Redis":
        return cls(decode_responses)

    def ping(self) -> bool:
        return True

    def execute_command(self, *args: Any) -> Any:
        raise ResponseError("unknown command (no modules in fake redis)")

    def hset(self, key: str, field: Optional[str] = None, value: Any = None, mapping=None) -> int:
        h = _DB.setdefault(str(key), {})
        if field is not None:
            h[_b(field)] = _b(value)
        for f, v in (mapping or {}).items():
            h[_b(f)] = _b(v)
        return 1

    def hget(self, key: str, field: str) -> Optional[bytes]:
        return _DB.get(str(key), {}).get(_b(field))

    def hgetall(self, key: str) -> Dict[bytes, bytes]:
        return dict(_DB.get(str(key), {}))

    def exists(self, key: str) -> int:
        return int(str(key) in _DB)

    def hexists(self, key: str, field: str) -> bool:
        return _b(field) in _DB.get(str(key), {})

    def delete(self, *keys: str) -> int:
        n = 0
        for k in keys:
            n += int(_DB.pop(str(k), None) is not None)
        return n

    def scan(self, cursor: int = 0, match: str = "*", count: int = 1000):
        keys = [k.encode("utf-8") for k in _DB.keys() if fnmatch.fnmatchcase(k, match)]
        return 0, keys

    def pipeline(self, transaction: bool = True) -> _Pipeline:
        return _Pipeline(self)


def reset() -> None:
    _DB.clear()


def install() -> None:
    mod = types.ModuleType("redis")
    mod.Redis = Redis
    mod.ResponseError = ResponseError
    mod.ConnectionError = ConnectionError
    sys.modules["redis"] = mod
