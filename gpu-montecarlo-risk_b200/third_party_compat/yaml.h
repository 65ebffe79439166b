/*
 * libyaml-subset compatibility header (build-time stand-in, written from scratch).
 *
 * libyaml 0.2.5 is a dependency of the reference's config loader
 * (src/utilities/load_config.c:5, Makefile:8 `-lyaml`) and is absent from this
 * image.  Only the event-parser surface that loader uses is declared
 * (SURVEY.md §8 row C): yaml_parser_{initialize,set_input_file,parse,delete},
 * yaml_event_delete, and the event types it switches on.  The parser handles
 * the block-style subset config.yaml is written in: nested block mappings,
 * block sequences of scalars, plain / single- / double-quoted scalars, comments.
 * It is NOT on the hot path; it exists so main_runner.c builds unmodified.
 */
#ifndef MCR_YAML_COMPAT_H
#define MCR_YAML_COMPAT_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char yaml_char_t;

typedef enum yaml_event_type_e {
    YAML_NO_EVENT,
    YAML_STREAM_START_EVENT,
    YAML_STREAM_END_EVENT,
    YAML_DOCUMENT_START_EVENT,
    YAML_DOCUMENT_END_EVENT,
    YAML_ALIAS_EVENT,
    YAML_SCALAR_EVENT,
    YAML_SEQUENCE_START_EVENT,
    YAML_SEQUENCE_END_EVENT,
    YAML_MAPPING_START_EVENT,
    YAML_MAPPING_END_EVENT
} yaml_event_type_t;

typedef struct yaml_event_s {
    yaml_event_type_t type;
    union {
        struct {
            yaml_char_t *anchor;
            yaml_char_t *tag;
            yaml_char_t *value;
            size_t length;
            int plain_implicit;
            int quoted_implicit;
            int style;
        } scalar;
    } data;
} yaml_event_t;

typedef struct yaml_parser_s {
    int error;
    FILE *input;
    /* pre-built event queue */
    struct mcr_yaml_ev *events;
    size_t n_events, cap_events, next_event;
    int built;
} yaml_parser_t;

int yaml_parser_initialize(yaml_parser_t *parser);
void yaml_parser_delete(yaml_parser_t *parser);
void yaml_parser_set_input_file(yaml_parser_t *parser, FILE *file);
int yaml_parser_parse(yaml_parser_t *parser, yaml_event_t *event);
void yaml_event_delete(yaml_event_t *event);

#ifdef __cplusplus
}
#endif
#endif
