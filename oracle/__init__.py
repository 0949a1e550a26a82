"""CPU oracle for the RVQ hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package. The product path
(``audio_generation_b200`` / ``som_quantizer``) never does.
"""
