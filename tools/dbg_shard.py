import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, torch
from graph_core_b200 import worlds as W
from graph_core_b200.sharded import CudaShardEngine
import oracle_py as orc
dim,n=12,3001
pts=W.c4_nodes(n,dim,seed=95); qs=W.c4_queries(4,dim,seed=96)
e=CudaShardEngine(dim,n+2,0); e.append(pts)
idx,dd,cnt=e.knn_lists(qs,5); torch.cuda.synchronize()
print('knn lists', idx.cpu().numpy(), cnt.cpu().numpy())
ov=orc.OracleNN('vector',dim); ov.insert(pts); print(ov.knn(qs,5,5)[0])
oi,od,oc=e.merge(idx.unsqueeze(0),dd.unsqueeze(0),cnt.unsqueeze(0),5,5); torch.cuda.synchronize()
print('merged', oi.cpu().numpy(), oc.cpu().numpy())
i1,d1,c1=e.nearest_lists(qs); print(i1.cpu().numpy().T, c1.cpu().numpy())
