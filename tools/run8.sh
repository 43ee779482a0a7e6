bash tools/run_ab.sh base "VF_X=0" bk64 "VF_TUNE_BK32_MAXK=0" wres "VF_TUNE_WRES=1" wres64 "VF_TUNE_WRES=1 VF_TUNE_BK32_MAXK=0" conv3 "VF_TUNE_CONV=3" conv15 "VF_TUNE_CONV=15" wres2 "VF_TUNE_WRES=2"
