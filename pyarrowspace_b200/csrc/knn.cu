// knn.cu -- K1, item orientation (nodes = items): eps / k-NN graph over the item rows.
// Placeholder until the item-graph kernel lands: fails loudly, never falls back.
#include "common.cuh"

int asp_item_knn(asp_space *s, const asp_graph_params *gp, asp_knn_lists *lists)
{
    (void)s; (void)gp; (void)lists;
    ASP_FAIL(ASP_ERR_UNSUPPORTED, "item-graph construction is not implemented yet");
}
