"""Drop-in import paths of the reference (`lib.net`, `lib.common`, `lib.dataset`) re-exporting icon_b200."""
