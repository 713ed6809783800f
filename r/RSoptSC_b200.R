# Drop-in R bodies for the hot path of RSoptSC, calling librsoptsc_b200.so through the .Call shim
# (r/rsoptsc_shim.c).  Signatures and return lists are the reference's
# (R/SimilarityM.R:27-28,96; R/Clustering.R:20-23,39-41,61,80-82,96-104).
# Not runnable in the build image (no R); see INTEGRATION.md.

SimilarityM <- function(lambda = 0.5, data, comps_knn = NULL,
                        use_umap_indices = FALSE, pre_embed_method = 'umap', ...){
  set.seed(1)                                           # R/SimilarityM.R:29
  data <- Matrix::as.matrix(data)
  storage.mode(data) <- "double"
  X2 <- NULL
  if (pre_embed_method == 'tsne'){                      # stochastic embeddings stay in R (:73-88)
    X <- data - min(data); X <- X / max(X)
    X <- sweep(X, 2, sqrt(colSums(X^2)), "/")
    X2 <- Rtsne::Rtsne(X = t(X), ...)$Y
  } else if (pre_embed_method == 'umap'){
    X2 <- umap::umap(d = t(data), n_components = if (is.null(comps_knn)) 2 else 3, ...)$layout
  }                                                     # 'pca': NULL -> in-library PCA embedding
  res <- .Call("rsoptsc_R_similarity", as.double(lambda), data, X2, 0L)
  list(W = res$W, E = res$E, nl_embedding = X2)
}

computeM <- function(D, X, lambda){
  .Call("rsoptsc_R_computeM", Matrix::as.matrix(D) * 1.0, Matrix::as.matrix(X) * 1.0, as.double(lambda))
}

GetComponents <- function(data, tol = 0.01){
  .Call("rsoptsc_R_get_components", Matrix::as.matrix(data) * 1.0, as.double(tol))
}

CountClusters <- function(data, tol = 0.01, range = 2:20, eigengap = TRUE, n_comp){
  .Call("rsoptsc_R_count_clusters", Matrix::as.matrix(data) * 1.0, as.double(tol),
        as.integer(min(range)), as.integer(max(range)), as.integer(n_comp))
}

ClusterCells <- function(similarityMatrix = NULL, n_clusters = NULL, n_comp = 3, ...){
  clusters <- NULL                                      # quirk Q11: undefined in the reference
  if (is.null(n_clusters)){
    clusters <- CountClusters(data = similarityMatrix, n_comp = n_comp)
    n_clusters <- clusters$upper_bound
  }
  res <- .Call("rsoptsc_R_nmf", Matrix::as.matrix(similarityMatrix) * 1.0, as.integer(n_clusters))
  list(H = res$H, labels = res$labels, ensemble = clusters)
}
