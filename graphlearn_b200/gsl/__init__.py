"""graphlearn_b200.gsl - see the package README / DESIGN.md for the layer map."""
