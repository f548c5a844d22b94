import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import adelie_b200 as ad
from oracle import oracle as orc
from cov_data import create_data_gaussian_pin_cov
args, ex = create_data_gaussian_pin_cov(10,100,20,13)
A = ad.matrix.dense(np.asfortranarray(ex["A"]), method="cov")
st = ad.state.gaussian_pin_cov(A=A, **args, tol=1e-12).solve()
print("rsqs", st.rsqs, "rdev_tol", st.rdev_tol, st.error)
print(np.diff(st.rsqs))
