class ItemItemRecommender:
    def __init__(self, K=20, num_threads=0):
        self.K = K
        self.num_threads = num_threads


class CosineRecommender(ItemItemRecommender):
    pass


class TFIDFRecommender(ItemItemRecommender):
    pass


class BM25Recommender(ItemItemRecommender):
    def __init__(self, K=20, K1=1.2, B=0.75, num_threads=0):
        super().__init__(K, num_threads)
        self.K1, self.B = K1, B
