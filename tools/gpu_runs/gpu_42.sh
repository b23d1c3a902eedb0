# probe: trace vs quadrature as the source of the I / S_b differences
python tools/bc_debug.py 2>&1 | tail -60
