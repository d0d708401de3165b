"""Known-answer scenarios transcribed (data only) from the reference's own tests; shared by the oracle
tests (CPU) and the parity tests (GPU).  Sources, relative to /root/reference/pkg/kvcache:
  kvblock_scorer_test.go:35-99   (scores 3.0 / 2.5 with weights gpu 1.0, cpu 0.5)
  indexer_test.go:121-234        (ScoreTokens scenarios with default weights gpu 1.0, cpu 0.8)
"""
POD_A, POD_B = "pod-a", "pod-b"

# (name, weights, keys, {key: [(pod, tier)]}, expected scores)
SCORER_CASES = [
    ("longest_prefix", {"gpu": 1.0, "cpu": 0.5}, [1001, 1002, 1003, 1004, 1005, 1006],
     {1001: [(POD_A, "gpu")], 1002: [(POD_A, "gpu")], 1003: [(POD_A, "gpu"), (POD_A, "cpu")],
      1004: [(POD_B, "cpu")], 1005: [(POD_B, "cpu")], 1006: [(POD_A, "gpu")]},
     {POD_A: 3.0}),
    ("different_tiers", {"gpu": 1.0, "cpu": 0.5}, [1001, 1002, 1003, 1004, 1005, 1006],
     {1001: [(POD_A, "gpu")], 1002: [(POD_A, "gpu")], 1003: [(POD_A, "cpu")],
      1004: [(POD_B, "cpu")], 1005: [(POD_B, "cpu")], 1006: [(POD_A, "gpu")]},
     {POD_A: 2.5}),
]

# (name, keys, index entries, pod filter, expected)   -- block size 1 so tokens == one block each
INDEXER_CASES = [
    ("no matching pods", [100, 200, 300], {}, [], {}),
    ("single pod full match", [10, 20, 30],
     {10: [(POD_A, "gpu")], 20: [(POD_A, "gpu")], 30: [(POD_A, "gpu")]}, [], {POD_A: 3.0}),
    ("multiple pods", [10, 20, 30],
     {10: [(POD_A, "gpu"), (POD_B, "gpu")], 20: [(POD_A, "gpu"), (POD_B, "gpu")], 30: [(POD_A, "gpu")]}, [],
     {POD_A: 3.0, POD_B: 2.0}),
    ("mixed device tiers", [10, 20], {10: [(POD_A, "gpu")], 20: [(POD_A, "cpu")]}, [], {POD_A: 1.8}),
    ("pod identifier filter", [10, 20],
     {10: [(POD_A, "gpu"), (POD_B, "gpu")], 20: [(POD_A, "gpu"), (POD_B, "gpu")]}, [POD_A], {POD_A: 2.0}),
    ("prefix break", [10, 20, 30],
     {10: [(POD_A, "gpu"), (POD_B, "gpu")], 20: [(POD_A, "gpu")], 30: [(POD_A, "gpu"), (POD_B, "gpu")]}, [],
     {POD_A: 3.0, POD_B: 1.0}),
    ("empty pod identifiers returns all", [10], {10: [(POD_A, "gpu"), (POD_B, "gpu")]}, [],
     {POD_A: 1.0, POD_B: 1.0}),
    ("deterministic", [10, 20], {10: [(POD_A, "gpu")], 20: [(POD_A, "gpu")]}, [], {POD_A: 2.0}),
]
