for occ in 0.2 0.25 0.3 0.35; do echo "== DC_KNN_OCC=$occ"; DC_KNN_OCC=$occ KNN_REC_VARIANTS=-1,5,7 timeout 250 python tools/check_knn_recorded.py 64; done
