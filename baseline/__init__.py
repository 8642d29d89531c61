"""Reference arm of bench.py: the reference's OWN Python (git-ignored copy under baseline/_ref/, made by
baseline/install_ref.py) over the reference's own CUDA kernels (oracle/_ref/pointnet2_ref_ext.so).

Nothing in this package imports scan2cap_b200 or the oracle restatements: the --impl reference process maps only
the reference extension."""
