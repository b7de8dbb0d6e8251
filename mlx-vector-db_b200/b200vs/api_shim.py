"""Thin HTTP shim over the B200 store with the reference's `/vectors/*` request and response
shapes (api/routes/vectors.py:163-417, service/models.py:34-61), so the reference's SDK and
integration test can talk to this engine.  Host-side glue only: no auth, rate limiting, admin or
monitoring routes (out of scope); every search goes through the C-ABI.

    from b200vs.api_shim import create_app
    app = create_app(base_path="~/.team_mind_data/vector_stores", dimension=384)

Differences from the reference, both deliberate:
* `/vectors/batch_query` works (the reference calls an undefined `store.batch_query`,
  api/routes/vectors.py:291) and is served by ONE batched GPU search;
* batch results use the same score convention as `/vectors/query` (cosine -> similarity,
  `distance = 1 - similarity`); the reference's batch route re-interprets the value as a
  distance (api/routes/vectors.py:303-304), which contradicts its own `/query` route.
"""
from __future__ import annotations

import re
import threading
import time
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
from fastapi import APIRouter, FastAPI, HTTPException
from pydantic import BaseModel, Field, model_validator

from .store import MLXVectorStore, MLXVectorStoreConfig


class VectorAddRequest(BaseModel):          # service/models.py:34-46
    user_id: str
    model_id: str
    vectors: List[List[float]]
    metadata: List[Dict[str, Any]]

    @model_validator(mode="after")
    def _same_length(self):
        if len(self.vectors) != len(self.metadata):
            raise ValueError("Vectors and metadata must have the same length")
        return self


class VectorQuery(BaseModel):               # service/models.py:48-54
    user_id: str
    model_id: str
    query: List[float]
    k: int = Field(default=10, ge=1, le=1000)
    filter_metadata: Optional[Dict[str, Any]] = None


class BatchQueryRequest(BaseModel):         # service/models.py:56-61
    user_id: str
    model_id: str
    queries: List[List[float]]
    k: int = 10


def _format(metric: str, raw_score: float, meta: Dict, rank: int) -> Dict[str, Any]:
    """api/routes/vectors.py:236-258."""
    if metric == "cosine":
        similarity, distance = raw_score, 1.0 - raw_score
    elif metric == "euclidean":
        distance = raw_score
        similarity = 1.0 / (1.0 + distance)
    else:
        similarity, distance = raw_score, -raw_score
    return {"metadata": meta, "similarity_score": float(similarity), "distance": float(distance), "rank": rank}


_ID_RE = re.compile(r"^[A-Za-z0-9][A-Za-z0-9.-]{0,63}$")


def _check_id(name: str, value: str) -> str:
    """The reference guards these routes with API keys (out of scope here); this shim at least never
    lets a request string leave `base_path`: ids are one path component of [A-Za-z0-9.-], no `_`
    (so the `user_model` key below is unambiguous), no `..`."""
    if not isinstance(value, str) or not _ID_RE.match(value) or ".." in value:
        raise HTTPException(status_code=400, detail=f"invalid {name}")
    return value


class StoreManager:
    """api/routes/vectors.py:37-144: one store per `user_model` key under
    `<base>/<user_id>/<model_id>` (:57), created on first ADD; at most `max_stores` live stores."""

    def __init__(self, base_path: str, dimension: int, metric: str, max_stores: int = 64, **config):
        self.base = Path(base_path).expanduser().resolve()
        self.dimension, self.metric, self.config = dimension, metric, config
        self.max_stores = max_stores
        self.stores: Dict[str, MLXVectorStore] = {}
        self.lock = threading.Lock()

    def get(self, user_id: str, model_id: str, create: bool = False) -> MLXVectorStore:
        """The store of (user, model).  create=False (queries, counts): an id pair nobody has added
        vectors for is a 404, not a new GPU handle -- unless a store of that name is on disk."""
        user_id, model_id = _check_id("user_id", user_id), _check_id("model_id", model_id)
        key = f"{user_id}_{model_id}"
        path = (self.base / user_id / model_id).resolve()
        if self.base not in path.parents:
            raise HTTPException(status_code=400, detail="invalid store path")
        with self.lock:
            st = self.stores.get(key)
            if st is None:
                if not create and not path.exists():
                    raise HTTPException(status_code=404, detail=f"Store not found: {user_id}/{model_id}")
                if len(self.stores) >= self.max_stores:
                    raise HTTPException(status_code=503, detail="too many open stores")
                cfg = MLXVectorStoreConfig(dimension=self.dimension, metric=self.metric, **self.config)
                st = self.stores[key] = MLXVectorStore(str(path), cfg)
            return st

    def close(self):
        with self.lock:
            for st in self.stores.values():
                st.close()
            self.stores.clear()


def create_app(base_path: str, dimension: int = 384, metric: str = "cosine", **config) -> FastAPI:
    app = FastAPI(title="b200vs /vectors shim")
    manager = StoreManager(base_path, dimension, metric, **config)
    app.state.store_manager = manager
    router = APIRouter(prefix="/vectors", tags=["vectors"])

    @router.post("/add")
    def add_vectors(request: VectorAddRequest):
        t0 = time.time()
        if not request.vectors or not request.metadata:
            raise HTTPException(status_code=400, detail="Vectors and metadata required")
        try:
            store = manager.get(request.user_id, request.model_id, create=True)
            store.add_vectors(np.asarray(request.vectors, dtype=np.float32), request.metadata)
        except HTTPException:
            raise
        except Exception as e:           # the reference maps everything else to 500 (:205-207)
            raise HTTPException(status_code=500, detail=f"Failed to add vectors: {e}")
        return {"success": True, "vectors_added": len(request.vectors),
                "total_vectors": store.get_stats()["vector_count"],
                "processing_time_ms": (time.time() - t0) * 1000}

    @router.post("/query")
    def query_vectors(request: VectorQuery):
        t0 = time.time()
        if not request.query:
            raise HTTPException(status_code=400, detail="Query vector required")
        try:
            store = manager.get(request.user_id, request.model_id)
            _, scores, metas = store.query(request.query, k=request.k, filter_metadata=request.filter_metadata)
        except HTTPException:
            raise
        except Exception as e:
            raise HTTPException(status_code=500, detail=f"Query failed: {e}")
        results = [_format(store.config.metric, s, m, i + 1) for i, (s, m) in enumerate(zip(scores, metas))]
        return {"results": results, "query_time_ms": (time.time() - t0) * 1000,
                "total_vectors_searched": store.get_stats()["vector_count"]}

    @router.post("/batch_query")
    def batch_query_vectors(request: BatchQueryRequest):
        t0 = time.time()
        if not request.queries:
            raise HTTPException(status_code=400, detail="Query vectors required")
        try:
            store = manager.get(request.user_id, request.model_id)
            batch = store.batch_query(np.asarray(request.queries, dtype=np.float32), k=request.k)
        except HTTPException:
            raise
        except Exception as e:
            raise HTTPException(status_code=500, detail=f"Batch query failed: {e}")
        results = [[_format(store.config.metric, s, m, i + 1) for i, (s, m) in enumerate(zip(scores, metas))]
                   for _, scores, metas in batch]
        total = (time.time() - t0) * 1000
        return {"results": results, "total_queries": len(request.queries),
                "avg_query_time_ms": total / len(request.queries)}

    @router.get("/count")
    def get_vector_count(user_id: str, model_id: str):
        store = manager.get(user_id, model_id)
        return {"count": store.get_stats()["vector_count"], "user_id": user_id, "model_id": model_id}

    app.include_router(router)
    return app
