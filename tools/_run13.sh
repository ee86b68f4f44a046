tools/order_sweep.sh "4" base
for v in stag2 stag4 stag8; do HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_$v.so tools/order_sweep.sh "4" $v; done
tools/order_sweep.sh "4" base
