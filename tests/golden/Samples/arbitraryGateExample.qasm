4
def1 Zee zee.gate
Zee 0
X 0
Y 0
Z 0
SWAP 0 1
CRk 1 2
def2 Cz cz.gate
Cz 0 3