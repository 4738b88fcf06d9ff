for v in quad tile; do
  DCVIC_FINISH=$v python tools/profile_run.py vq D0 6 > gpurun_out/p27_$v.log 2>&1 && DCVIC_FINISH=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors.sum,l1tex__t_sector_hit_rate.pct --clock-control none -s 12 -c 6 --csv --log-file gpurun_out/launches27_$v.csv python tools/profile_run.py vq D0 6 > /dev/null 2>&1
done
