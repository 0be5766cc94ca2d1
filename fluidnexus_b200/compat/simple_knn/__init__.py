"""Drop-in for the reference's `simple_knn` package (FD/submodules/simple-knn); see simple_knn._C."""
