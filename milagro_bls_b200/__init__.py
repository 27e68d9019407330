"""milagro_bls_b200 -- B200-native (sm_100a) BLS12-381 batch-verification engine behind the milagro_bls
verification API.  See DESIGN.md / INTEGRATION.md.  The CUDA library is mandatory: there is no CPU fallback."""
from ._lib import LIB_PATH, EXPORTED_SYMBOLS, B3LibraryMissing  # noqa: F401
from .api import (AggregatePublicKey, AggregateSignature, AmclError, Comm, Engine, KeyTable, PublicKey, Signature,  # noqa: F401
                  default_engine, nccl_unique_id, set_default_engine)
from .rng import SeededRng, draw_scalar  # noqa: F401

__all__ = ["AggregatePublicKey", "AggregateSignature", "AmclError", "Comm", "Engine", "KeyTable", "PublicKey", "Signature", "SeededRng",
           "nccl_unique_id",
           "draw_scalar", "default_engine", "set_default_engine", "LIB_PATH", "EXPORTED_SYMBOLS"]
