import sys, os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from golden_lib import SCENE_KW, load, state
from gpu_helpers import context_from_oracles
from oracle_lib import Oracle
for scene in sorted(SCENE_KW):
    g=load(scene); o=Oracle(scene,**SCENE_KW[scene]); ctx=context_from_oracles([o],max_contacts=1024)
    ctx.step(float(g["dt"]),int(g["total"])); ctx.sync()
    gs,fin=ctx.download_state(),state(g,"final")
    print(scene,int(g["total"]),{k:float(np.abs(gs[k][0]-fin[k]).max()) for k in "xqvw"})
    ctx.close()
