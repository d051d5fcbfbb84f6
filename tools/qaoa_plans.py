import os, sys, json
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"]="1"; os.environ["QTB_QAOA_VERBOSE"]="1"
from qtorch_b200 import host_api
G=os.path.join(ROOT,'tests','golden')
rec=json.load(open(G+'/maxcut.json'))['3reg30_p2_default']
for r in (18, 11, 43, 2):
    q=host_api.QaoaObjective(os.path.join(G,rec['graph']),2,rank=r,world=45)
    q.close()
