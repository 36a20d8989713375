import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
from helpers import *
from bayesgp_b200 import make_objective
model = covid_model()[0]
ff = make_objective(tmbdata_from_oracle(model))
theta=np.array([0.0]); W=np.zeros(model.p)
o=model.objective(W,theta,'fgH')
f,g,H=ff.objective(W,theta,True,True)
E=H-o['H']
print('max abs err', np.abs(E).max(), 'at', np.unravel_index(np.abs(E).argmax(), E.shape))
np.set_printoptions(linewidth=250, precision=2)
print((np.abs(E)/np.maximum(1e-300,np.abs(o['H']))).max())
R=np.abs(E)/np.maximum(1e-300,np.abs(o['H']))
print('relative err per row max:', R.max(1))
print('abs err matrix corner'); print(E[:8,:8]); print(E[30:,30:])
print('nonzero err count', (np.abs(E)>1e-9*np.abs(o['H']).max()).sum())
print('f err', f - o['f'], 'g relerr', np.abs(g - o['g']).max() / np.abs(o['g']).max())
bad = np.abs(E) > 1e-9 * np.abs(o['H']).max()
print('bad rows', np.where(bad.any(1))[0]); print('bad cols', np.where(bad.any(0))[0])
