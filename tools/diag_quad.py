import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import nmpc_b200 as gpu, oracle_lib as O
from test_quadrotor_gpu import quadrotor_x0, hover_inputs, N
np.set_printoptions(linewidth=200, precision=6)
B=8192; sub=1024; p=O.default_params("quadrotor"); x0,u0=quadrotor_x0(B,4),hover_inputs(B)
for name in ("quadrotor","quadrotor_f64"):
    s=gpu.DDPSolver(name,params=p,batch_capacity=B); c=s.config(); c.horizon_steps,c.max_iter,c.k_rel_norm_thre,c.cost_update_thre=N,10,0.0,0.0
    s.solve_batch(0.0,x0,u0); cost=s.cost()
    ref=O.ddp_solve_batch("quadrotor",p,O.ddp_config(max_iter=10,horizon_steps=N,k_rel_norm_thre=0.0,cost_update_thre=0.0),x0[:sub],u0[:sub])
    rel=np.abs(cost[:sub]-ref["cost"])/np.abs(ref["cost"])
    print(name,"max",rel.max(),"q99",np.quantile(rel,0.99),"frac>1e-3",(rel>1e-3).mean(),"frac>1e-5",(rel>1e-5).mean())
    tr=s.trace()
    for b in np.argsort(-rel)[:3]:
        print(" b",b,"rel",rel[b],"x0",x0[b,:6])
        print("  gpu cost",tr[b,:,1]); print("  ref cost",ref["trace"][b,:,1]); print("  gpu alpha",tr[b,:,4]); print("  ref alpha",ref["trace"][b,:,4]); print("  gpu lambda",tr[b,:,2]); print("  ref lambda",ref["trace"][b,:,2])
