set -x
python profiles/profile_case.py 6000 3e-3 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_fit' -s 2 -c 2 -o /tmp/prof_fit python profiles/profile_case.py 6000 3e-3 > gpurun_out/ncu.log 2>&1
ls -la /tmp/prof_fit.ncu-rep
ncu -i /tmp/prof_fit.ncu-rep --page raw --csv > gpurun_out/prof_fit_raw.csv 2>/dev/null
ncu -i /tmp/prof_fit.ncu-rep --page source --csv --print-source sass > /tmp/prof_fit_source.csv 2>/dev/null
gzip -c /tmp/prof_fit_source.csv > gpurun_out/prof_fit_source.csv.gz
ncu -i /tmp/prof_fit.ncu-rep --page details > gpurun_out/prof_fit_details.txt 2>/dev/null
cp /tmp/prof_fit.ncu-rep gpurun_out/ 2>/dev/null; du -sh gpurun_out
tail -3 gpurun_out/prof_plain.log
