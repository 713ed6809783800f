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
  IDX <- NULL
  if (pre_embed_method == 'tsne'){                      # stochastic embeddings stay in R (:73-88)
    X <- data - min(data); X <- X / max(X)
    X <- sweep(X, 2, sqrt(colSums(X^2)), "/")
    X2 <- Rtsne::Rtsne(X = t(X), ...)$Y
  } else if (pre_embed_method == 'umap'){
    # exactly the reference's call (:79): comps_knn is handed to umap as is.  With the default NULL, umap falls
    # back to its own n_components = 2 and the reference then fails at X2[, 1:3] (:83, quirk Q2); the library
    # fails the same way ("embedding needs 3 columns") rather than silently using 2
    data.umap <- umap::umap(d = t(data), n_components = comps_knn, ...)
    X2 <- data.umap$layout
    if (ncol(X2) < 3) stop("subscript out of bounds: the embedding has fewer than 3 columns (set comps_knn >= 3)")
    if (use_umap_indices){                              # :85-87
      IDX <- data.umap$knn$indexes
      storage.mode(IDX) <- "integer"
    }
  }                                                     # 'pca': NULL -> in-library PCA embedding
  res <- .Call("rsoptsc_R_similarity", as.double(lambda), data, X2, 0L, IDX)
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
  dots <- list(...)                                     # `...` goes to NMF::nmf in the reference (:29-33)
  max_iter <- if (!is.null(dots$maxIter)) dots$maxIter else if (!is.null(dots$.options$maxIter)) dots$.options$maxIter else 2000L
  unsupported <- setdiff(names(dots), c("maxIter", ".options"))
  if (length(unsupported)) warning("ClusterCells: NMF arguments ignored by the GPU path: ", paste(unsupported, collapse = ", "))
  res <- .Call("rsoptsc_R_nmf", Matrix::as.matrix(similarityMatrix) * 1.0, as.integer(n_clusters), as.integer(max_iter))
  list(H = res$H, labels = res$labels, ensemble = clusters)
}

GetEnsemble <- function(data, tol, n_comp = 3, tau, range = 2:20){   # R/Clustering.R:125-154
  .Call("rsoptsc_R_get_ensemble", Matrix::as.matrix(data) * 1.0, as.integer(n_comp), as.double(tau),
        as.integer(min(range)), as.integer(max(range)))
}

GetConsensus <- function(clusters){                                   # R/Clustering.R:164-175
  .Call("rsoptsc_R_get_consensus", as.integer(clusters))
}

# R/DataSelection.R:117-133, :63-76, :91-103, :17-51
LogNormalize <- function(M, scale = 10000, normalize = TRUE){
  .Call("rsoptsc_R_log_normalize", Matrix::as.matrix(M) * 1.0, as.double(scale), as.integer(normalize))
}
ScaleCenterData <- function(M){
  .Call("rsoptsc_R_scale_center", Matrix::as.matrix(M) * 1.0)
}
ApplyCountThreshold <- function(M, countL = 1, countU = 3, X = 3){
  .Call("rsoptsc_R_count_threshold", Matrix::as.matrix(M) * 1.0, as.double(countL), as.double(countU), as.double(X))
}
