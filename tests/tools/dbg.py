import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cases as CS
from oracle import oracle
from spark_tfrecord_b200 import _native
c=[x for x in CS.reference_cases() if x.name=='ref_sequence_example'][0]
data=c.data()
want=oracle.decode(data,c.schema,1)
dec=_native.Decoder(c.schema,1)
b,used=dec.decode(data)
cols=b.to_host()
print(b.info)
for f,g,w in zip(c.schema,cols,want.columns):
    print(f.name, "valid", g.valid(0), w.valid(0), "cnt", [list(o) for o in g.offsets], [list(o) for o in w.offsets])
