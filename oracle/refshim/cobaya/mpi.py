def is_main_process():
    return True


def root_only(func):
    return func
