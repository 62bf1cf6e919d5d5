import sys
sys.path.insert(0, ".")
import numpy as np
from scanpy_b200 import _ops
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rs = np.random.RandomState(n + d)
x = rs.standard_normal((n, d)).astype(np.float32); x[: n // 3] += 2.5
idx, dist, info = _ops.knn(x, k)
print(info)
