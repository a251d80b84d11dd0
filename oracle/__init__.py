"""CPU oracle for the FrAD hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of what the reference
(H4n-uL/FrAD_Python, ``/root/reference/src/libfrad``) computes on the hot path.
It exists to check the CUDA product path and to serve as the timed CPU baseline
in ``bench.py``.  Nothing in ``frad_python_b200/`` may import it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do.

Pinning status (see DESIGN.md "Oracle"):
  * DCT / profile 0 / 1 / 2 / CRC restatements are pinned against outputs of the
    reference itself, generated in the build container by ``oracle/make_golden.py``
    and committed under ``tests/golden/``.
  * Reed-Solomon: the reference delegates to the third-party ``reedsolo`` package,
    which is absent from the image.  ``oracle/rs.py`` restates its
    published algorithm and is pinned only by public known-answer vectors; the
    reference holds no vector of its own => "parity unpinned by the reference".
"""
