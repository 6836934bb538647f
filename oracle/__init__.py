"""CPU oracle for the hetero-GNN + heads hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package; the product (``multi-modal-art-classifier_b200/``) never does.

PARITY STATUS
  * GNN half (ToUndirected, to_hetero, SAGEConv/GraphConv, propagate): **parity unpinned** against
    PyG itself.  PyG 2.0.2 / torch-scatter 2.0.9 (environment.yml:225,244 of the reference) are not
    vendored under /root/reference and not installable here, and the reference holds no tests or
    golden vectors.  The restatement follows PyG 2.0.2's published algorithm with the same ATen CPU
    kernels (index_select -> scatter_add_ -> F.linear).  What IS pinned: the network wiring.  The
    reference's own ``src/models/models_graph.py`` is imported unmodified in this container (with
    ``torch_geometric.nn`` bound to this oracle's operators) by ``tests/golden/make_golden.py`` and
    its outputs are committed as fixtures.
  * Heads half (projector, new-multimodal fusion heads, losses): pinned.  The arithmetic of the
    reference's ``src/models/models_kg.py`` classes is first-party torch; ``make_golden.py`` imports
    those classes unmodified (backbones stubbed by an identity feature extractor) and commits their
    outputs, losses and gradients as fixtures.
"""
