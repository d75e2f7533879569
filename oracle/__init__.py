"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the SegCLIP dual-encoder forward/backward hot path
(reference: ArrowLuo/SegCLIP, modules/modeling.py:174-256 and callees).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker.
The product (``segclip_b200``) never imports it and has no CPU fallback.
"""
