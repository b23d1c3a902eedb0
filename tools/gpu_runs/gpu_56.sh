# second invariant on the device (GuidingCenter.geteye), full GPU suite
python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
