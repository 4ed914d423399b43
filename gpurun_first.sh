set -x
make -s -C oracle
python -m pytest tests -q -m gpu 2>&1 | tail -25
python __graft_entry__.py --smoke
