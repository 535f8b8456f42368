# After tools/final_r02.sh: the summaries it wrote on the GPU box (gpurun_out/prof_$tag/) -> profiles/, plus the traffic table
# and the SASS record of the shipped library (run here, in the build container).
tag=${1:-r02}
cp gpurun_out/prof_$tag/* profiles/
python tools/sass_hist.py > profiles/sass_$tag.txt
python tools/traffic_r02.py $tag
