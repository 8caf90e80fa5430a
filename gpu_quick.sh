python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
python tools/build_one.py greenland_1km >> gpurun_out/t.log 2>&1
python tools/build_one.py antarctica_1km >> gpurun_out/t.log 2>&1
