make -s -C oracle
python tests/parity_at_scale.py gpurun_out/parity_r1.json 2>&1 | grep -o '"ok": [a-z]*.*"config": "C[0-9]"' | cut -c1-80
