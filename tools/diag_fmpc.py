import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import nmpc_b200 as gpu, oracle_lib as O
B,N=1024,100
x0=O.cartpole_x0(B,3); p=O.default_params("fmpc_cartpole")
for mi in (2,4,6,8,10):
    solver=gpu.FmpcSolver("cartpole",params=p,batch_capacity=B); solver.config().max_iter=mi
    var=solver.make_variable(B); var.reset(0,0,0,1,1)
    st=solver.solve_batch(0.0,x0,var)
    d={"x":var.x_list,"u":var.u_list,"lambda":var.lambda_list,"s":var.s_list,"nu":var.nu_list}
    ref=O.fmpc_solve_batch("fmpc_cartpole",p,O.fmpc_config(horizon_steps=N,max_iter=mi),x0,d)
    v=solver.variable(); out={"x":v.x_list,"u":v.u_list,"lambda":v.lambda_list,"s":v.s_list,"nu":v.nu_list}
    err=np.zeros(B)
    for k in out:
        ax=tuple(range(1,out[k].ndim))
        err=np.maximum(err,np.max(np.abs(out[k]-ref[k]),axis=ax)/(1+np.max(np.abs(ref[k]),axis=ax)))
    print(mi,"status eq",np.array_equal(st,ref["status"]),"frac<=1e-8",(err<=1e-8).mean(),"max",err.max(),"p99",np.quantile(err,0.99), "kkt final med", np.median(ref["trace"][:,mi-1,1]))
