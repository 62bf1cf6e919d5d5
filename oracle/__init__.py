"""CPU oracle for the pca -> neighbors -> leiden hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from this package; `scanpy_b200` (the product) never does.

What each module is, and how it is pinned to the reference (scverse/scanpy @ fabadb94):

* `oracle.pca`      - the reference's own arithmetic: scanpy's `pca()` body is one call to
                      `sklearn.decomposition.PCA(svd_solver='arpack')`
                      (src/scanpy/preprocessing/_pca/__init__.py:282-291,308); sklearn/scipy are
                      installed, so this IS the reference run here.  Pinned by the reference's
                      golden A_list -> A_pca (tests/test_pca.py:34-59,225-233).
* `oracle.knn`      - the reference's exact kNN path: `KNeighborsTransformer(algorithm='brute')`
                      (src/scanpy/neighbors/__init__.py:754-768) plus a restatement of
                      src/scanpy/neighbors/_common.py:35-98.  Pinned by the 4-point golden
                      (tests/test_neighbors.py:23-39) and the in-tree pbmc68k fixture (700/700 rows).
* `oracle.fuzzy`    - numpy restatement of umap-learn's `fuzzy_simplicial_set`
                      (umap-learn >= 0.5.12, NOT vendored in /root/reference and not installed;
                      call site src/scanpy/neighbors/_connectivity.py:124-138).  Pinned by the
                      golden `connectivities_umap` (tests/test_neighbors.py:43-48) and by the
                      fixture's stored distances -> connectivities.
* `oracle.leiden`   - C restatement (oracle/leiden_ref.c) of the Leiden algorithm as run by
                      leidenalg.find_partition(RBConfigurationVertexPartition) (leidenalg >= 0.10.1,
                      NOT vendored, not installed; call site src/scanpy/tools/_leiden.py:184-187).
                      Label level: the reference's tests hold no Leiden label golden
                      (tests/test_clustering.py pins properties only); the restatement is checked against
                      those properties, Traag et al.'s guarantees, planted partitions, networkx modularity,
                      and the one clustering the reference itself wrote to disk - the in-tree fixture's
                      `obs/louvain` (sc.tl.louvain defaults), reproduced at ARI 0.94-0.98 on the unweighted
                      graph (tests/test_oracle_leiden_guarantees.py).
"""
