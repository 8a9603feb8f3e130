"""Generate tests/golden/cache_golden.json from the REFERENCE's own ``CachedSRDataset`` (run in the build container):

    PYTHONPATH=/root/reference python oracle/make_cache_golden.py

The mock cache is rebuilt from a seed by ``oracle.cache_oracle.write_mock_cache`` (torch CPU generator), so only the
digests of what the reference class returns are committed, not the tensors.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cache_oracle as CO  # noqa: E402


def main():
    from src.data.cached_dataset import CachedSRDataset          # the reference class (needs PYTHONPATH=/root/reference)
    with tempfile.TemporaryDirectory() as d:
        CO.write_mock_cache(d, **CO.GOLDEN_MOCK)
        with contextlib.redirect_stdout(io.StringIO()):
            res = CO.golden_digests(CachedSRDataset, d)
    out = os.path.join(ROOT, "tests", "golden", "cache_golden.json")
    with open(out, "w") as f:
        json.dump({"source": "src/data/cached_dataset.py CachedSRDataset, mock cache " + repr(CO.GOLDEN_MOCK), "cases": res}, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
