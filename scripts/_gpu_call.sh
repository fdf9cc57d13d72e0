python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | cut -c1-200
