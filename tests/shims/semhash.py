"""placeholder for the optional `semhash` dependency, which ref:llmq/workers/__init__.py:5 imports
unconditionally (semantic dedup worker — out of scope). TEST SHIM ONLY."""


class SemHash:  # pragma: no cover
    @classmethod
    def from_records(cls, *a, **k):
        raise RuntimeError("semhash is not installed in this environment")
