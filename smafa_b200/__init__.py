"""smafa_b200 -- B200 (sm_100a) drop-in for the query/cluster hot path of wwood/smafa.

The product is libsmafa_b200.so (CUDA kernels behind the C ABI of include/smafa_b200.h) plus
the `smafa` CLI.  This package is the thin ctypes binding used by the tests, bench.py and the
multi-GPU driver; importing it never touches the CPU oracle.
"""
from .api import (Context, Db, SmafaError, SmafaPanic, cluster, count, lib_path, load_db_file, load_library, makedb,  # noqa: F401
                  query)
