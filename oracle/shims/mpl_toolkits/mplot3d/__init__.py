from matplotlib import _Inert

Axes3D = _Inert
