"""Host-side mirror of the reference's `pointnet2/` directory (same module names and public
symbols), backed by libeda_b200.so.  Put this directory's PARENT on sys.path as `pointnet2`'s
provider or import `eda_b200.pointnet2.*` directly."""
