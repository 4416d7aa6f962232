set -e
cp gstools-core_b200/gstools_core/libgsfield.so /tmp/hybrid.so
echo "== hybrid build"; GSF_TAIL_WAVES=0 python tools/tail_probe.py; GSF_TAIL_WAVES=0.5 python tools/tail_probe.py
cp gpurun_scratch/libgsfield_nohybrid.so gstools-core_b200/gstools_core/libgsfield.so
echo "== no-hybrid build"; GSF_TAIL_WAVES=0 python tools/tail_probe.py
cp /tmp/hybrid.so gstools-core_b200/gstools_core/libgsfield.so
