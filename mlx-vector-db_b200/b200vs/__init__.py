"""b200vs -- B200-native exact vector search (drop-in for the reference's store hot path).

    from b200vs import MLXVectorStore, MLXVectorStoreConfig, create_optimized_vector_store
    from b200vs import ops            # performance/mlx_optimized.py surface on torch tensors
"""
from .store import MLXVectorStore, MLXVectorStoreConfig, create_optimized_vector_store  # noqa: F401

__all__ = ["MLXVectorStore", "MLXVectorStoreConfig", "create_optimized_vector_store"]
