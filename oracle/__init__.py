"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the HandsOnVLM visual-token path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU baseline), never as
something the CUDA path falls back to.

* ``synth.py``    -- seeded synthetic weights / inputs.
* ``restate.py``  -- fp32 torch-CPU restatement of every stage of the path, each
                     function citing the reference file:line it follows.
* ``ref_shim.py`` -- import shim that loads the *real* reference modules from
                     ``/root/reference`` (dev container only; the reference does not
                     travel to the GPU box).
* ``make_golden.py`` -- runs the real reference (+ HF ``CLIPVisionModel``) here and
                     freezes small fixtures into ``tests/golden/``; the restatement is
                     pinned against those fixtures by ``tests/test_oracle_*.py``.

Parity status: the reference ships **no tests or golden vectors** (SURVEY.md section 4), so the
oracle is pinned against *outputs of the reference itself run in the dev container*
(fixtures + generating script committed).
"""
