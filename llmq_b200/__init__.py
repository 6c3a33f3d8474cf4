"""llmq_b200 — a B200-native (sm_100a) batched-generation worker for llmq's worker slot.

Importing the package never touches CUDA; `llmq_b200.lib.load()` loads libb200q.so and raises if
it has not been built (there is no CPU fallback)."""
__version__ = "0.1.0"
