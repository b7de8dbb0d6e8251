"""Import-path shim: with `mlx-vector-db_b200/` ahead of the reference on sys.path,
`from service.optimized_vector_store import MLXVectorStore` (as api/routes/vectors.py:29 and
integrations/mlx_lm_pipeline.py:50 of the reference do) resolves to the B200 engine."""
from b200vs.store import (MLXVectorStore, MLXVectorStoreConfig,  # noqa: F401
                          create_optimized_vector_store)
