#!/usr/bin/env python
"""CPU timing of `rectools_b200.recommend()` AROUND an instantaneous ranker at bench scale (the `model_recommend` leg of bench.py
minus the engine): unmodified reference model + Dataset from oracle/_ref, ~100 viewed items per user.

    python scripts/host_profile.py 1000000 1000000
"""
import os, sys, time
import numpy as np
ROOT=os.path.abspath(os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, ROOT)
from oracle import stage_reference
stage_reference.add_to_path()
import pandas as pd
from rectools import Columns
from rectools.dataset import Dataset, IdMap, Interactions
import rectools_b200
import importlib; R=importlib.import_module("rectools_b200.recommend"); I=importlib.import_module("rectools_b200.integration")
from tests.ref_models import injected_als
import bench
n_users=int(sys.argv[1]); n_items=int(sys.argv[2]); per=100; K=10; d=128
users=bench.gen_factors(n_users,d,1); items=bench.gen_factors(n_items,d,0)
indptr,indices=bench.gen_viewed(n_users,n_items,per)
rows=np.repeat(np.arange(n_users,dtype=np.int64),np.diff(indptr))
keep=np.ones(len(indices),bool); keep[1:]=(indices[1:]!=indices[:-1])|(rows[1:]!=rows[:-1])
df=pd.DataFrame({Columns.User:rows[keep],Columns.Item:indices[keep].astype(np.int64)}); del rows,keep
df[Columns.Weight]=np.float64(1.0); df[Columns.Datetime]=pd.Timestamp("2024-01-01")
dataset=Dataset(IdMap(np.arange(n_users,dtype=np.int64)),IdMap(np.arange(n_items,dtype=np.int64)),Interactions(df))
model=injected_als(users,items)
rng=np.random.default_rng(0)
IDS=rng.integers(0,n_items,(n_users,K)).astype(np.int32); SC=np.sort(rng.random((n_users,K),dtype=np.float32),axis=1)[:,::-1].copy()
class Fake:
    def __init__(self,dist,u,i,**kw):
        self.distance="dot"
        t=time.perf_counter(); self.h=(I.content_hash(np.asarray(u)),I.content_hash(np.asarray(i))); print("  factor hashes %.3f"%(time.perf_counter()-t))
    def rank_padded(self,s,k=None,filter_pairs_csr=None,sorted_object_whitelist=None,flags=0):
        s=np.asarray(s); return s,IDS,SC,np.full(len(s),K,np.int32)
all_users=dataset.user_id_map.external_ids
import cProfile,pstats
for it in range(3):
    t=time.perf_counter(); out=R.recommend(model,all_users,dataset,K,True,ranker_factory=Fake); print("recommend %.3f s"%(time.perf_counter()-t), len(out))
pr=cProfile.Profile(); pr.enable(); out=R.recommend(model,all_users,dataset,K,True,ranker_factory=Fake); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
