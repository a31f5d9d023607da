import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quickbench as q
N=(256,256,128)
for fo in ('slot','cell'):
    os.environ['MINIAERO_FACE_ORDER']=fo
    for gv,fv,tile,bt in [('tma','tma',(8,4,4),0),('tma','tma',(4,4,4),0),('tma','tma',(8,8,4),0),('tma','tma',(4,4,8),0),('tma','tma',(8,4,4),128),('tma','gather',(8,8,4),0)]:
        q.variant(gv,fv)
        try:
            q.run(*N,1,1,tile=tile,bt=bt,tag=fo+' '+gv+'/'+fv)
        except Exception as e:
            print('FAILED',fo,gv,fv,tile,bt,e,flush=True)
os.environ['MINIAERO_FACE_ORDER']='slot'
q.variant('tma','tma')
for exp in (1,2,3):
    os.environ['MINIAERO_EXP']=str(exp)
    q.run(*N,1,1,tile=(8,4,4),tag='exp%d'%exp)
