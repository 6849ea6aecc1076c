"""CPU oracle for the OpenScene hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``openscene_b200/``
or ``MinkowskiEngine/`` (the product) imports it; the product fails loudly when its
CUDA library is missing instead of falling back to this code.

Parity status ("Oracle"):

* ``voxelize_ref``  -- restates ``dataset/voxelizer.py`` + ``dataset/voxelization_utils.py``
  and is PINNED: ``tests/golden/voxelizer_*.npz`` were produced by importing the
  reference's own ``Voxelizer`` in the build container
  (``scripts/make_golden.py``) and the restatement is checked against them bit-exactly.
* ``me_cpu``        -- restates the MinkowskiEngine semantics the reference relies on.
  MinkowskiEngine is an un-vendored, un-pinned git-HEAD dependency of the reference
  (``installation.md:37-39``) that is neither in ``/root/reference`` nor installable
  offline, and the reference ships no tests or golden vectors for it, so this part is
  **parity unpinned**: it is an independent restatement of the published generalised
  sparse convolution (Choy et al., CVPR'19) anchored on the reference's call sites
  (``models/mink_unet.py``, ``models/resnet_base.py``).  The network *topology* is pinned:
  the golden activations were produced by running the reference's unmodified
  ``models/mink_unet.py`` on top of ``me_cpu``.
* ``matching``      -- restates ``run/evaluate.py:283-326`` and ``run/distill.py:322-328``.
* ``fusion_ref``    -- restates ``PointCloudToImageMapper.compute_mapping`` (PINNED to the reference's own outputs,
  ``tests/golden/fusion_mapping_*.npz``) and the accumulate loop of ``scripts/feature_fusion/scannet_openseg.py``.
* ``loader_ref``    -- restates the fused-feature remap of ``dataset/feature_loader.py:101-172`` (PINNED to the
  outputs of the reference's own ``FusedFeatureLoader``, ``tests/golden/loader_*.npz``).
* ``metric_ref``    -- restates ``util/metric.py`` and ``util/util.py:117-145`` (PINNED, ``tests/golden/metric_*.npz``).
"""
