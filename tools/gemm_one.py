import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tools")
import gemm_bench
M, N, K, bn = (int(a) for a in sys.argv[1:5])
gemm_bench.run(M, N, K, bn, reps=5)
