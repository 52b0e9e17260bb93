"""oracle/ — CPU restatement of the VRP-Gym rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under this package is shipped or measured as
the product: it may be imported solely by `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py`, and there only
as the checker (or as the CPU baseline that is *reported beside* the GPU path).

Parity status:
  * env transitions, instance stream, RandomAgent: PINNED — checked against the
    6,912 Random-Agent golden costs of reference `reproduction_log/*.csv`, the
    known answers of reference `tests/test_env.py`, `tests/test_agent.py`, and
    transition tapes recorded from the unmodified reference
    (`tests/golden/make_golden.py`).
  * encoder / decoder / greedy rollout: PINNED against logits, embeddings and
    greedy tapes recorded from the unmodified reference with seed-initialised
    weights (`tests/golden/policy_*.npz`) and the reference's own test means
    (tests/test_agent.py:84,99,114).
  * trained-checkpoint parity: UNPINNED (the .pt blobs are absent from the
    reference tree, `.MISSING_LARGE_BLOBS`).
  * sampled action streams: UNPINNED (reference `test_decoder` no longer
    reproduces on torch 2.11); teacher-forced log-probs are pinned instead.
"""
