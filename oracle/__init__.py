"""CPU oracle for the COSKAD anomaly-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``coskad_b200/`` imports this package.  The only
callers allowed are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker / the timed CPU
baseline, never as the product path.

What it restates (reference paths are relative to the upstream COSKAD tree):

* ``oracle.stsgcn``      -- models/graph_layers/stsgcn.py, models/common/components.py,
                            models/sts/ae.py, models/sts/vae.py  (torch CPU, same ATen ops)
* ``oracle.geoopt_math`` -- geoopt==0.5.0 ``manifolds/stereographic/math.py`` with k=-1
                            (third-party, pinned in environment.yml:247, NOT in the tree and
                            not installable offline  ->  **parity unpinned** for this module)
* ``oracle.hyper_math``  -- utils/hyper_math.py flavour (c=+1 convention); pinned against
                            the real module by tests/golden/geometry_hyper_math.npz
* ``oracle.power_spherical`` -- nicola-decao/power_spherical (unpinned, un-vendored ->
                            **parity unpinned**)
* ``oracle.aggregate``   -- utils/eval_utils.py:57-106 + eval_COSKAD.py:140-253 in numpy
* ``oracle.mahalanobis`` -- utils/eval_utils.py:28-55 + models/euclidean_encoder_staticCenter.py:40-46,
                            133-142; pinned by tests/golden/mahalanobis_ref.npz (the reference's functions)
The VAE network part of ``oracle.stsgcn`` (``stsvae_encode``) is pinned by the reference's own
``models/sts/vae.py`` STSVAE (tests/golden/stsvae_ref.npz); only the PowerSpherical sampling stays unpinned.

Pinning status: the network restatement is checked against the real reference modules
(imported from /root/reference in the build container) by ``oracle/gen_golden.py``; the
outputs are committed under ``tests/golden/``.  The reference ships no tests or golden
vectors of its own.
"""
