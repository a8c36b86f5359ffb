import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import os
import nmpc_b200 as gpu, oracle_lib as O
from test_quadrotor_gpu import quadrotor_x0, hover_inputs, N
B=64; p=O.default_params("quadrotor"); x0,u0=quadrotor_x0(B,9),hover_inputs(B)
lo=np.array([7.0,-0.05,-0.05,-0.02]); hi=np.array([12.0,0.05,0.05,0.02])
for mi in (1,2,3,6):
    ref=O.ddp_solve_batch("quadrotor",p,O.ddp_config(max_iter=mi,horizon_steps=N,with_input_constraint=1),x0,u0,u_lo=lo,u_hi=hi)
    s=gpu.DDPSolver("quadrotor_f64",params=p,batch_capacity=B); c=s.config(); c.horizon_steps,c.max_iter,c.with_input_constraint=N,mi,True
    s.setInputLimitsFunc((lo,hi)); s.solve_batch(0.0,x0,u0)
    u=s.controlData().u_list
    rel=np.max(np.abs(u-ref["u"]),axis=(1,2))/(1+np.max(np.abs(ref["u"]),axis=(1,2)))
    k=s.k_list(); 
    relk=np.max(np.abs(k-ref["k"]),axis=(1,2))/(1+np.max(np.abs(ref["k"]),axis=(1,2)))
    print("max_iter",mi,"GS",os.environ.get("NMPC_B200_BWD_GS"),"rel_u q50 %.2e q90 %.2e max %.2e"%(np.median(rel),np.quantile(rel,.9),rel.max()),"rel_k max %.2e"%relk.max(), "iters eq", np.array_equal(s.iterations(),ref["iters"]), "nfwd eq", np.array_equal(s.n_forward(), ref["n_fwd"]), "cost rel max %.2e"%np.max(np.abs(s.cost()-ref["cost"])/np.abs(ref["cost"])))
mi=3
ref=O.ddp_solve_batch("quadrotor",p,O.ddp_config(max_iter=mi,horizon_steps=N,with_input_constraint=1),x0,u0,u_lo=lo,u_hi=hi)
s=gpu.DDPSolver("quadrotor_f64",params=p,batch_capacity=B); c=s.config(); c.horizon_steps,c.max_iter,c.with_input_constraint=N,mi,True
s.setInputLimitsFunc((lo,hi)); s.solve_batch(0.0,x0,u0)
u=s.controlData().u_list
rel=np.max(np.abs(u-ref["u"]),axis=(1,2))/(1+np.max(np.abs(ref["u"]),axis=(1,2)))
np.set_printoptions(linewidth=220, precision=12)
print("n_bwd eq", np.array_equal(s.n_backward(), ref["n_bwd"]), s.n_backward()[:16], ref["n_bwd"][:16])
tr=s.trace()
for b in np.argsort(-rel)[:3]:
    print("instance",b,"rel",rel[b])
    print(" gpu trace cost/lambda/alpha/k_rel", tr[b,:,1], tr[b,:,2], tr[b,:,4], tr[b,:,5])
    print(" ref trace cost/lambda/alpha/k_rel", ref["trace"][b,:,1], ref["trace"][b,:,2], ref["trace"][b,:,4], ref["trace"][b,:,5])
    k=s.k_list()[b]; dk=np.abs(k-ref["k"][b]); i=np.unravel_index(np.argmax(dk), dk.shape); print(" worst k diff at step",i, k[i[0]], ref["k"][b][i[0]])
    K=s.K_list()[b]; print(" K gpu", K[i[0]]); print(" K ref", ref["K"][b][i[0]].reshape(12,4).T)
