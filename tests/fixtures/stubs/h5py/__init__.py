"""Test harness ONLY: an empty stand-in for the `h5py` package, which is absent from this image.

The reference scripts `import h5py` at the top (examples/test.py:7) although the evaluation path never calls it; the
drop-in test puts this directory on PYTHONPATH so that the UNMODIFIED script imports -- also inside the DataLoader
worker processes, which re-import the main script under the 'spawn' start method.  The product's PCA store detects that
this module has no `File` attribute and writes its parameters to '<path>.npz' instead (openibl_b200/pca.py)."""
