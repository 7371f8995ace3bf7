def non_neg():
    return None
