import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad
ad.set_configs("sweep_ctas", int(os.environ.get("CTAS", 0)))
data = ad.data.dense(int(os.environ.get("N", 1000)), 50, 10, seed=3)
st = ad.grpnet(data["X"], data["glm"], groups=data["groups"], progress_bar=False, lmda_path_size=5, early_exit=False)
print("err", st.error, len(st.lmdas), st.devs)
